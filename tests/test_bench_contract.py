"""bench.py contract checks that run without a GPU: the reference arm prints one JSON line with the agreed keys
(it times the C restatement of the reference on the host cores), and the product arm refuses to run without a
CUDA device instead of falling back to the CPU."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=300, env=env, cwd=ROOT)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--particles", "3000", "--cpu-sample", "3000")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "particles/s" and line["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--particles", "1000", "--cpu-sample", "1000"], capture_output=True, text=True,
                       timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    r = _run("--steps", "1", "--warmup", "3", "--particles", "1000", "--no-cpu")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_product_never_touches_the_oracle():
    """The oracle is the checker: no file of the product (rubix_b200/, Python or CUDA / C++) may import, link or name
    it; in bench.py it appears only inside the CPU legs (cpu_arm) and the parity object."""
    import re
    hits = []
    for base, _, files in os.walk(os.path.join(ROOT, "rubix_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", "Makefile")):
                text = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"\boracle\b|librubix_oracle|rbxo(32|64)_", text):
                    hits.append(os.path.join(base, f))
    assert hits == [], hits
    # the shipped library does not link the checker either
    so = os.path.join(ROOT, "rubix_b200", "librubix_b200.so")
    if os.path.exists(so):
        out = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
        assert "oracle" not in out
