"""bench.py's reference arm runs on host cores only, so its JSON contract can be checked without a GPU: one line, the
metric / config keys of the b200 arm, `"impl": "reference"`, a cpu_baseline that says which CPU implementation ran, and
an e2e object with zero transfer bytes.  Under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--impl", "reference", "--steps", "2", "--warmup", "1", "--width", "96", "--height", "72", "--samples", "4"]


def run(extra_env=None, args=ARGS):
    env = dict(os.environ, **(extra_env or {}))
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    return [json.loads(line) for line in res.stdout.splitlines() if line.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = run()
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "Mpaths/s" and d["unit"] == "Mpaths/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "reference scene 96x72@4spp" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_uses_the_compiled_reference_kernel_when_it_is_built():
    from oracle import oracle as O
    d = run()[0]
    assert d["cpu_baseline"]["kind"] == ("reference" if O.ref_lib() is not None else "port")


def test_reference_arm_is_silent_on_other_ranks():
    assert run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []


def test_reference_arm_on_a_mesh_scene():
    d = run(args=["--impl", "reference", "--steps", "1", "--warmup", "0", "--scene", "teapot", "--aperture", "0", "--focal-length", "0",
                  "--width", "64", "--height", "48", "--samples", "2"])[0]
    assert d["value"] > 0 and "teapot" in d["config"]["workload"]


def test_config_table_matches_baseline_json():
    """bench.py --config N must be BASELINE.json configs[N-1]: sizes and sample counts are checked against its text."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    configs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    for key, idx in (("1", 0), ("2", 1), ("3", 2), ("4", 3), ("5", 4), ("5e", 4), ("5c", 4)):
        scene, w, h, spp, ap, fl, label = bench.CONFIGS[key]
        text = configs[idx]
        assert f"{w}x{h}" in text and (f"{spp} spp" in text or f"@{spp}" in text.replace(" ", "") or f"at {spp} spp" in text), (key, text)
        assert f"configs[{idx}]" in label
    assert bench.CONFIGS["2"][4:6] == (0.15, 1.6) and "aperture 0.15" in configs[1] and "focal length 1.6" in configs[1]
    assert bench.workload_name("reference", 1280, 960, 2048, 0.15, 1.6, bench.CONFIGS["2"][6]).startswith("reference scene 1280x960@2048spp")
